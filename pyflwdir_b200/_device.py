"""DeviceGraph: one D8 raster resident on one GPU (thin object wrapper over the C ABI handle)."""
import ctypes as C

import numpy as np

from . import _lib

_IDX_DTYPES = (np.dtype(np.int32), np.dtype(np.uint32), np.dtype(np.int64), np.dtype(np.uint64))


def nodata_args(nodata):
    """(float64 value, int64 value, is_int) -- numba compares int data with an int nodata as int64 and
    everything else as float64 (pyflwdir/streams.py:39 under numba typing)."""
    is_int = isinstance(nodata, (int, np.integer)) and not isinstance(nodata, (bool, np.bool_))
    nd_f = float(nodata)
    nd_i = int(nodata) if is_int else 0
    if is_int and not (-(2**63) <= nd_i < 2**63):
        is_int, nd_i = False, 0
    return C.c_double(nd_f), C.c_int64(nd_i), C.c_int(1 if is_int else 0)


class DeviceGraph:
    def __init__(self, device=0):
        self._l = _lib.lib()
        h = C.c_void_p()
        _lib.check(self._l.pfd_create(int(device), C.byref(h)))
        self._h = h
        self.device = int(device)
        # PFD_TILES=0 selects the level-synchronous BFS + sweeps for rank / basins / upstream_area("cell")
        # instead of the tile-hierarchical solver (identical results; used by the parity tests)
        import os
        if os.environ.get("PFD_TILES", "1") == "0":
            self.set_option("tiles", 0)
        # accuflux / Strahler / HAND engine: default 1 = tile-dataflow sweeps unless the BFS ordering is already cached;
        # PFD_TILE_SWEEPS=2 always tile-dataflow, =0 always level replays over the BFS order (identical results; used by
        # the parity tests)
        if os.environ.get("PFD_TILE_SWEEPS", "1") in ("0", "2"):
            self.set_option("tile_sweeps", int(os.environ["PFD_TILE_SWEEPS"]))
        # HAND: PFD_HAND_PATHSUM=0 skips the re-associated path sums (pfd_hand.cuh) and goes straight to the hop-by-hop sweeps
        if os.environ.get("PFD_HAND_PATHSUM", "1") == "0":
            self.set_option("hand_pathsum", 0)
        self.shape = None
        self.size = 0
        self.n_valid = self.n_pits = self.n_outlets = 0
        self.nnodes = None
        self.nlevels = None

    # -- lifetime
    def close(self):
        if getattr(self, "_h", None) is not None:
            self._l.pfd_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, status):
        _lib.check(status, self._h)

    def set_option(self, name, value):
        self._ck(self._l.pfd_set_option(self._h, name.encode(), int(value)))

    def info(self, name):
        return int(self._l.pfd_get_info(self._h, name.encode()))

    # -- the step before the path: dem.fill_depressions (does not touch the parsed raster of this handle)
    def fill_depressions(self, elevtn, outlets="edge", idxs_pit=None, nodata=-9999.0, max_depth=-1.0, elv_max=None,
                         connectivity=8):
        e = np.asarray(elevtn)
        if e.ndim != 2:
            raise ValueError("elevtn should be a 2D array")
        int_delv = 0
        if e.dtype in (np.float32, np.float64):
            w = np.ascontiguousarray(e)
        elif np.issubdtype(e.dtype, np.integer) or e.dtype == np.bool_:
            w, int_delv = np.ascontiguousarray(e, dtype=np.float64), 1  # numba's typing of the loop for integer rasters
        else:
            w = np.ascontiguousarray(e, dtype=np.float64)
        if outlets not in ("edge", "min"):
            outlets = "edge"  # the reference only tests `outlets == "min"`
        mode = 2 if idxs_pit is not None else (1 if outlets == "min" else 0)
        pits = np.ascontiguousarray([] if idxs_pit is None else idxs_pit, dtype=np.int64).ravel()
        out = _lib.out_array(w.size, w.dtype)
        d8 = _lib.out_array(w.size, np.uint8)
        stats = np.zeros(12, dtype=np.int64)
        self._ck(self._l.pfd_fill_depressions(
            self._h, _lib.ptr(w), _lib.dtype_code(w.dtype), w.shape[0], w.shape[1], mode, _lib.ptr(pits) if pits.size else None,
            pits.size, C.c_double(float(nodata)), C.c_double(float(max_depth)), int(elv_max is not None and mode != 2),
            C.c_double(0.0 if elv_max is None else float(elv_max)), int(connectivity), int_delv, _lib.ptr(out), _lib.ptr(d8),
            _lib.ptr(stats)))
        self.fill_stats = dict(zip(("level_passes", "label_passes", "tied_cells", "tie_components", "unreached", "outlets", "band",
                                    "largest_component", "max_drift", "tries", "levels_us", "ties_us"), stats.tolist()))
        for k in ("band", "max_drift"):  # float32 bit patterns
            self.fill_stats[k] = float(np.array([self.fill_stats[k]], dtype=np.uint32).view(np.float32)[0])
        out = np.array(out).reshape(w.shape)
        if int_delv:
            out = out.astype(e.dtype)
        return out, np.array(d8).reshape(w.shape)

    # -- parse
    def parse_d8(self, d8, idx_dtype=None, want_idxs=False, ftype="d8"):
        """core_d8.from_array / core_ldd.from_array on the device. 2-D uint8 raster (host) -> optional idxs_ds."""
        d8 = np.ascontiguousarray(d8, dtype=np.uint8)
        nrow, ncol = d8.shape
        nv, npit, nout = C.c_int64(), C.c_int64(), C.c_int64()
        idxs = None
        code = 0
        if want_idxs:
            idxs = _lib.out_array(d8.size, idx_dtype)
            code = _lib.dtype_code(idx_dtype)
        if ftype == "ldd":
            self._ck(self._l.pfd_ldd_parse(self._h, _lib.ptr(d8), nrow, ncol, _lib.ptr(idxs), code, C.byref(nv),
                                           C.byref(npit)))
        else:
            self._ck(self._l.pfd_d8_parse(self._h, _lib.ptr(d8), nrow, ncol, 1, _lib.ptr(idxs), code,
                                          C.byref(nv), C.byref(npit), C.byref(nout)))
        self._set_shape(nrow, ncol, nv.value, npit.value, nout.value)
        return idxs

    def parse_nextxy(self, nextx, nexty, idx_dtype=None, want_idxs=False, check=True):
        """core_nextxy.from_array on the device. Two 2-D int32 planes (host) -> optional idxs_ds."""
        nextx = np.ascontiguousarray(nextx, dtype=np.int32)
        nexty = np.ascontiguousarray(nexty, dtype=np.int32)
        if nextx.ndim != 2 or nextx.shape != nexty.shape:
            raise ValueError("NEXTXY planes must be two 2-D arrays of the same shape")
        nrow, ncol = nextx.shape
        nv, npit, nout = C.c_int64(), C.c_int64(), C.c_int64()
        idxs, code = None, 0
        if want_idxs:
            idxs = _lib.out_array(nextx.size, idx_dtype)
            code = _lib.dtype_code(idx_dtype)
        self._ck(self._l.pfd_nextxy_parse(self._h, _lib.ptr(nextx), _lib.ptr(nexty), nrow, ncol, 1 if check else 0, _lib.ptr(idxs),
                                          code, C.byref(nv), C.byref(npit), C.byref(nout)))
        self._set_shape(nrow, ncol, nv.value, npit.value, nout.value)
        return idxs

    def flow_all(self, d8, idx_dtype=np.int32, resident=True):
        """pfd_d8_flow_all: parse + rank + upstream_area("cell") + basins() in one call. With `resident` the raster
        and the four outputs live in device buffers for the call (the fused-parse path of the library); otherwise
        host buffers are handed to the library. Returns (idxs_ds, rank, uparea, basins) as flat host arrays."""
        d8 = np.ascontiguousarray(d8, dtype=np.uint8)
        nrow, ncol = d8.shape
        n = d8.size
        idx_dtype = np.dtype(idx_dtype)
        outs = [_lib.out_array(n, idx_dtype), _lib.out_array(n, np.int32), _lib.out_array(n, np.int32),
                _lib.out_array(n, np.uint32)]
        nv, npit, nn = C.c_int64(), C.c_int64(), C.c_int64()
        if resident:
            bufs = []
            try:
                for nbytes in [n] + [o.nbytes for o in outs]:
                    p = C.c_void_p()
                    self._ck(self._l.pfd_dev_alloc(self._h, max(int(nbytes), 16), C.byref(p)))
                    bufs.append(p)
                self._ck(self._l.pfd_memcpy(self._h, bufs[0], _lib.ptr(d8), n))
                self._ck(self._l.pfd_d8_flow_all(self._h, bufs[0], nrow, ncol, bufs[1], _lib.dtype_code(idx_dtype), bufs[2],
                                                 bufs[3], bufs[4], C.byref(nv), C.byref(npit), C.byref(nn)))
                for o, b in zip(outs, bufs[1:]):
                    self._ck(self._l.pfd_memcpy(self._h, _lib.ptr(o), b, o.nbytes))
            finally:
                for b in bufs:
                    self._l.pfd_dev_free(self._h, b)
        else:
            self._ck(self._l.pfd_d8_flow_all(self._h, _lib.ptr(d8), nrow, ncol, _lib.ptr(outs[0]), _lib.dtype_code(idx_dtype),
                                             _lib.ptr(outs[1]), _lib.ptr(outs[2]), _lib.ptr(outs[3]), C.byref(nv),
                                             C.byref(npit), C.byref(nn)))
        self._set_shape(nrow, ncol, nv.value, npit.value, 0)
        self.n_outlets = self.info("n_outlets")
        return tuple(outs)

    def load_idxs_ds(self, idxs_ds, shape):
        idxs_ds = np.ascontiguousarray(idxs_ds)
        if idxs_ds.dtype not in _IDX_DTYPES:
            raise TypeError(f"idxs_ds dtype {idxs_ds.dtype} not supported")
        nv, npit = C.c_int64(), C.c_int64()
        self._ck(self._l.pfd_load_idxs_ds(self._h, _lib.ptr(idxs_ds), _lib.dtype_code(idxs_ds.dtype), int(shape[0]),
                                          int(shape[1]), C.byref(nv), C.byref(npit)))
        self._set_shape(int(shape[0]), int(shape[1]), nv.value, npit.value, 0)

    def _set_shape(self, nrow, ncol, nv, npit, nout):
        self.shape = (nrow, ncol)
        self.size = nrow * ncol
        self.n_valid, self.n_pits, self.n_outlets = nv, npit, nout
        self.nnodes = self.nlevels = None

    # -- order
    def order(self):
        nn, nl = C.c_int64(), C.c_int64()
        self._ck(self._l.pfd_order(self._h, C.byref(nn), C.byref(nl)))
        self.nnodes, self.nlevels = nn.value, nl.value
        return self.nnodes, self.nlevels

    def fetch(self, which, idx_dtype=np.int32):
        code = 0
        if which == _lib.ARR_IDXS_DS:
            out = _lib.out_array(self.size, idx_dtype)
            code = _lib.dtype_code(idx_dtype)
        elif which == _lib.ARR_PITS:
            out = np.empty(self.n_pits, dtype=idx_dtype)
            code = _lib.dtype_code(idx_dtype)
        elif which == _lib.ARR_PIT_IS_OUTLET:
            out = np.empty(self.n_pits, dtype=np.uint8)
        elif which == _lib.ARR_SEQ:
            if self.nnodes is None:
                self.order()
            out = _lib.out_array(self.nnodes, idx_dtype)
            code = _lib.dtype_code(idx_dtype)
        elif which == _lib.ARR_RANK:
            out = _lib.out_array(self.size, np.int32)
        elif which == _lib.ARR_N_UPSTREAM:
            out = _lib.out_array(self.size, np.int8)
        elif which in (_lib.ARR_D8, _lib.ARR_LDD):
            out = _lib.out_array(self.size, np.uint8)
        elif which == _lib.ARR_NEXTXY:
            out = _lib.out_array(2 * self.size, np.int32)
        elif which == _lib.ARR_LEVEL_OFFSETS:
            if self.nlevels is None:
                self.order()
            out = np.empty(self.nlevels + 1, dtype=np.int64)
        else:
            raise ValueError(which)
        if out.size or which in (_lib.ARR_SEQ, _lib.ARR_LEVEL_OFFSETS):
            self._ck(self._l.pfd_fetch(self._h, which, _lib.ptr(out) if out.size else _lib.ptr(np.empty(1, out.dtype)), code))
        if which in (_lib.ARR_SEQ, _lib.ARR_RANK) and self.nnodes is None:
            self.order()
        return out

    # -- sweeps (flat host arrays in, flat host arrays out)
    def accuflux(self, data, nodata, direction="up"):
        data = np.ascontiguousarray(data)
        if data.size != self.size:
            raise ValueError('"data" size does not match.')
        dt = data.dtype
        if dt == np.bool_:
            raise TypeError("accuflux: boolean data is not supported")
        out = _lib.out_array(data.size, dt)
        nd_f, nd_i, nd_is = nodata_args(nodata)
        self._ck(self._l.pfd_accuflux(self._h, _lib.ptr(data), _lib.dtype_code(dt), nd_f, nd_i, nd_is,
                                      0 if direction == "up" else 1, _lib.ptr(out)))
        return out

    def upstream_sum(self, data, nodata):
        """arithmetics.upstream_sum: sum of the values of the direct upstream neighbours."""
        data = np.ascontiguousarray(data)
        if data.size != self.size:
            raise ValueError('"data" size does not match.')
        if data.dtype == np.bool_:
            raise TypeError("upstream_sum: boolean data is not supported")
        out = _lib.out_array(data.size, data.dtype)
        nd_f, nd_i, nd_is = nodata_args(nodata)
        self._ck(self._l.pfd_upstream_sum(self._h, _lib.ptr(data), _lib.dtype_code(data.dtype), nd_f, nd_i, nd_is,
                                          _lib.ptr(out)))
        return out

    def subbasins_streamorder(self, strord, mask=None, min_sto=-2, idx_dtype=np.int32):
        """basins.subbasins_streamorder -> (int32 map, outlet indices)."""
        strord = np.ascontiguousarray(strord, dtype=np.uint8)
        if strord.size != self.size:
            raise ValueError('"strord" size does not match.')
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask)
            m = m.view(np.uint8) if m.dtype == np.bool_ else (m != 0).view(np.uint8)
            if m.size != self.size:
                raise ValueError('"mask" size does not match.')
        out = _lib.out_array(self.size, np.int32)
        k = C.c_int64()
        self._ck(self._l.pfd_subbasins_streamorder(self._h, _lib.ptr(strord), _lib.ptr(m), int(min_sto), _lib.ptr(out),
                                                   C.byref(k)))
        idxs = np.empty(k.value, dtype=idx_dtype)
        if k.value:
            self._ck(self._l.pfd_fetch(self._h, _lib.ARR_SUBBASIN_OUTLETS, _lib.ptr(idxs), _lib.dtype_code(idx_dtype)))
        return out, idxs

    def subbasins_area(self, idxs_us_main, uparea, area_min, idx_dtype=np.int32):
        """basins.subbasins_area -> (uint32 map, outlet indices)."""
        um = np.ascontiguousarray(idxs_us_main)
        if um.dtype not in _IDX_DTYPES or um.size != self.size:
            raise ValueError('"idxs_us_main" must be an index array of the raster size')
        upa = np.ascontiguousarray(uparea)
        if upa.size != self.size:
            raise ValueError('"uparea" size does not match.')
        if upa.dtype not in (np.dtype(np.int32), np.dtype(np.int64), np.dtype(np.float32), np.dtype(np.float64)):
            upa = upa.astype(np.float64)
        out = _lib.out_array(self.size, np.uint32)
        k = C.c_int64()
        self._ck(self._l.pfd_subbasins_area(self._h, _lib.ptr(um), _lib.dtype_code(um.dtype), _lib.ptr(upa),
                                            _lib.dtype_code(upa.dtype), C.c_double(float(area_min)), _lib.ptr(out), C.byref(k)))
        idxs = np.empty(k.value, dtype=idx_dtype)
        if k.value:
            self._ck(self._l.pfd_fetch(self._h, _lib.ARR_SUBBASIN_OUTLETS, _lib.ptr(idxs), _lib.dtype_code(idx_dtype)))
        return out, idxs

    def subbasins_pfafstetter(self, idxs_us_main, uparea, mask=None, depth=1, idx_dtype=np.int32):
        """basins.subbasins_pfafstetter -> (int64 map, outlet indices)."""
        um = np.ascontiguousarray(idxs_us_main)
        if um.dtype not in _IDX_DTYPES or um.size != self.size:
            raise ValueError('"idxs_us_main" must be an index array of the raster size')
        upa = np.ascontiguousarray(uparea)
        if upa.size != self.size:
            raise ValueError('"uparea" size does not match.')
        if upa.dtype not in (np.dtype(np.int32), np.dtype(np.int64), np.dtype(np.float32), np.dtype(np.float64)):
            upa = upa.astype(np.float64)
        out = _lib.out_array(self.size, np.int64)
        k = C.c_int64()
        self._ck(self._l.pfd_subbasins_pfafstetter(self._h, _lib.ptr(um), _lib.dtype_code(um.dtype), _lib.ptr(upa),
                                                   _lib.dtype_code(upa.dtype), _lib.ptr(self._mask_u8(mask, "mask")), int(depth),
                                                   _lib.ptr(out), C.byref(k)))
        return out, self._fetch_selected(k.value, idx_dtype)

    def classify_estuary(self, est_init, rivdst, rivwth, min_convergence):
        """rivers.classify_estuary -> int8 map."""
        arrs = []
        for a, name in ((rivdst, "rivdst"), (rivwth, "rivwth")):
            a = np.ascontiguousarray(a)
            if a.size != self.size:
                raise ValueError(f'"{name}" size does not match.')
            if a.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
                a = a.astype(np.float64)
            arrs.append(a)
        est_init = np.ascontiguousarray(est_init, dtype=np.int8)
        out = _lib.out_array(self.size, np.int8)
        self._ck(self._l.pfd_classify_estuary(self._h, _lib.ptr(est_init), _lib.ptr(arrs[0]), _lib.dtype_code(arrs[0].dtype),
                                              _lib.ptr(arrs[1]), _lib.dtype_code(arrs[1].dtype), C.c_double(float(min_convergence)),
                                              _lib.ptr(out)))
        return out

    def _window_args(self, data, idxs_us_main, strord):
        data = np.ascontiguousarray(data)
        if data.size != self.size:
            raise ValueError('"data" size does not match.')
        if data.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            raise TypeError("moving window: data must be float32 or float64")
        um = np.ascontiguousarray(idxs_us_main)
        if um.dtype not in _IDX_DTYPES or um.size != self.size:
            raise ValueError('"idxs_us_main" must be an index array of the raster size')
        so = None
        if strord is not None:
            so = np.ascontiguousarray(strord, dtype=np.uint8)
            if so.size != self.size:
                raise ValueError('"strord" size does not match.')
        return data, um, so

    def moving_average(self, data, weights, n, idxs_us_main, strord=None, nodata=-9999.0):
        """arithmetics.moving_average"""
        data, um, so = self._window_args(data, idxs_us_main, strord)
        w, wcode = None, 0
        if weights is not None:
            w = np.ascontiguousarray(weights)
            if w.size != self.size:
                raise ValueError('"weights" size does not match.')
            # arithmetics.py:101 only types with float64 weights (np.ones() in the other branch): anything else is upcast
            if w.dtype != np.dtype(np.float64):
                w = w.astype(np.float64)
            wcode = _lib.dtype_code(w.dtype)
        out = _lib.out_array(data.size, data.dtype)
        self._ck(self._l.pfd_moving_average(self._h, _lib.ptr(data), _lib.dtype_code(data.dtype), _lib.ptr(w), wcode, int(n),
                                            _lib.ptr(um), _lib.dtype_code(um.dtype), _lib.ptr(so), C.c_double(float(nodata)),
                                            _lib.ptr(out)))
        return out

    def moving_median(self, data, n, idxs_us_main, strord=None, nodata=-9999.0):
        """arithmetics.moving_median"""
        data, um, so = self._window_args(data, idxs_us_main, strord)
        out = _lib.out_array(data.size, data.dtype)
        self._ck(self._l.pfd_moving_median(self._h, _lib.ptr(data), _lib.dtype_code(data.dtype), int(n), _lib.ptr(um),
                                           _lib.dtype_code(um.dtype), _lib.ptr(so), C.c_double(float(nodata)), _lib.ptr(out)))
        return out

    # -- local traces and region post-processing
    def _mask_u8(self, a, name):
        if a is None:
            return None
        a = np.ascontiguousarray(a)
        if a.size != self.size:
            raise ValueError(f'"{name}" size does not match.')
        # any nonzero value is True (like numpy's truth value in the reference), not just values that survive a uint8 cast
        return (a != 0).view(np.uint8) if a.dtype != np.bool_ else a.view(np.uint8)

    def downstream(self, data):
        """Flwdir.downstream: value of the next downstream cell."""
        data = np.ascontiguousarray(data)
        if data.size != self.size:
            raise ValueError('"data" size does not match.')
        out = _lib.out_array(self.size, data.dtype)
        self._ck(self._l.pfd_downstream(self._h, _lib.ptr(data), _lib.dtype_code(data.dtype), _lib.ptr(out)))
        return out

    def trace(self, idxs0, direction="down", idxs_us_main=None, mask=None, max_length=None, hop_table=None, paths_dtype=None):
        """core.path (paths_dtype given) / core.snap (paths_dtype None) -> (paths or None, ends int64, dists float64)."""
        starts = np.ascontiguousarray(idxs0, dtype=np.int64).ravel()
        n0 = starts.size
        up = 1 if direction == "up" else 0
        um, ucode = None, 0
        if up:
            um = np.ascontiguousarray(idxs_us_main)
            if um.dtype not in _IDX_DTYPES or um.size != self.size:
                raise ValueError('"idxs_us_main" must be an index array of the raster size')
            ucode = _lib.dtype_code(um.dtype)
        m = self._mask_u8(mask, "mask")
        hop = None if hop_table is None else np.ascontiguousarray(hop_table, dtype=np.float64)
        if hop is not None and hop.size != self.shape[0] * 6:
            raise ValueError("hop table must hold nrow * 3 * 2 lengths")
        counts = np.zeros(n0, dtype=np.int64)
        ends = np.zeros(n0, dtype=np.int64)
        dists = np.zeros(n0, dtype=np.float64)
        has_max, mx = (0, 0.0) if max_length is None else (1, float(max_length))

        def call(paths, pcode, cap):
            self._ck(self._l.pfd_trace(self._h, _lib.ptr(starts), n0, up, _lib.ptr(um), ucode, _lib.ptr(m), has_max, C.c_double(mx),
                                       _lib.ptr(hop), _lib.ptr(counts), _lib.ptr(ends), _lib.ptr(dists), _lib.ptr(paths), pcode, cap))

        call(None, 0, 0)
        if paths_dtype is None:
            return None, ends, dists
        total = int(counts.sum())
        flat = np.empty(max(total, 1), dtype=paths_dtype)
        if n0:
            call(flat, _lib.dtype_code(paths_dtype), total)
        offs = np.concatenate([[0], np.cumsum(counts)])
        return [flat[offs[i]:offs[i + 1]].copy() for i in range(n0)], ends, dists

    def _fetch_selected(self, k, idx_dtype):
        idxs = np.empty(k, dtype=idx_dtype)
        if k:
            self._ck(self._l.pfd_fetch(self._h, _lib.ARR_SUBBASIN_OUTLETS, _lib.ptr(idxs), _lib.dtype_code(idx_dtype)))
        return idxs

    def inflow_idxs(self, region, idx_dtype=np.int32):
        k = C.c_int64()
        self._ck(self._l.pfd_inflow_idxs(self._h, _lib.ptr(self._mask_u8(region, "region")), C.byref(k)))
        return self._fetch_selected(k.value, idx_dtype)

    def outflow_idxs(self, region, idx_dtype=np.int32):
        k = C.c_int64()
        self._ck(self._l.pfd_outflow_idxs(self._h, _lib.ptr(self._mask_u8(region, "region")), C.byref(k)))
        return self._fetch_selected(k.value, idx_dtype)

    def interbasin_mask(self, region, stream=None):
        out = _lib.out_array(self.size, np.uint8)
        self._ck(self._l.pfd_interbasin_mask(self._h, _lib.ptr(self._mask_u8(region, "region")),
                                             _lib.ptr(self._mask_u8(stream, "stream")), _lib.ptr(out)))
        return out.view(np.bool_)

    @staticmethod
    def _region_array(regions):
        reg = np.ascontiguousarray(regions)
        if reg.dtype.kind not in "iu":
            raise TypeError("regions must be an integer array")
        if reg.dtype.itemsize < 4:  # small integer labels are widened for the device
            reg = reg.astype(np.int32)
        return reg

    def region_outlets(self, regions, idx_dtype=np.int32):
        """regions.region_outlets -> (labels in the dtype of regions, outlet cells)."""
        rdt = np.asarray(regions).dtype
        reg = self._region_array(regions)
        if reg.size != self.size:
            raise ValueError('"regions" size does not match.')
        k = C.c_int64()
        self._ck(self._l.pfd_region_outlets(self._h, _lib.ptr(reg), _lib.dtype_code(reg.dtype), C.byref(k)))
        lbs = np.empty(k.value, dtype=np.int64)
        if k.value:
            self._ck(self._l.pfd_fetch(self._h, _lib.ARR_REGION_LABELS, _lib.ptr(lbs), 0))
        return lbs.astype(rdt), self._fetch_selected(k.value, idx_dtype)

    def region_slices(self, regions):
        """regions.region_slices -> (ascending labels in the dtype of regions, int32 [n, 4] row / column slices)."""
        rdt = np.asarray(regions).dtype
        reg = self._region_array(regions)
        if reg.size != self.size:
            raise ValueError('"regions" size does not match.')
        k = C.c_int64()
        self._ck(self._l.pfd_region_slices(self._h, _lib.ptr(reg), _lib.dtype_code(reg.dtype), C.byref(k)))
        lbs = np.empty(k.value, dtype=np.int64)
        sl = np.empty((k.value, 4), dtype=np.int32)
        if k.value:
            self._ck(self._l.pfd_fetch(self._h, _lib.ARR_REGION_LABELS, _lib.ptr(lbs), 0))
            self._ck(self._l.pfd_fetch(self._h, _lib.ARR_REGION_SLICES, _lib.ptr(sl), 0))
        return lbs.astype(rdt), sl

    def streams(self, mask=None, max_len=0, idx_dtype=np.int32):
        """streams.streams -> list of index arrays (one per stream segment)."""
        ns, nc = C.c_int64(), C.c_int64()
        self._ck(self._l.pfd_streams(self._h, _lib.ptr(self._mask_u8(mask, "mask")), int(max_len), C.byref(ns), C.byref(nc)))
        offs = np.empty(ns.value + 1, dtype=np.int64)
        self._ck(self._l.pfd_fetch(self._h, _lib.ARR_STREAM_OFFSETS, _lib.ptr(offs), 0))
        cells = np.empty(max(nc.value, 1), dtype=idx_dtype)
        if nc.value:
            self._ck(self._l.pfd_fetch(self._h, _lib.ARR_STREAM_CELLS, _lib.ptr(cells), _lib.dtype_code(idx_dtype)))
        return np.split(cells[: nc.value], offs[1:-1]) if ns.value else []

    def upstream_area_cells(self):
        out = _lib.out_array(self.size, np.int32)
        self._ck(self._l.pfd_upstream_area_cells(self._h, _lib.ptr(out)))
        return out

    def basins(self, idxs=None, ids=None):
        if idxs is None:
            out = _lib.out_array(self.size, np.uint32)
            self._ck(self._l.pfd_basins(self._h, None, 0, 0, None, 0, _lib.ptr(out)))
            return out
        idxs = np.ascontiguousarray(idxs)
        if idxs.dtype not in _IDX_DTYPES:
            idxs = idxs.astype(np.int64)
        ids = np.ascontiguousarray(ids)
        if ids.dtype.kind not in "iu":
            raise TypeError("basin ids must be integers")
        # numpy fancy assignment keeps the LAST id of duplicated outlets (basins.py:17): dedupe on the host
        if idxs.size > 1:
            norm = np.where(idxs.astype(np.int64) < 0, idxs.astype(np.int64) + self.size, idxs.astype(np.int64))
            _, last = np.unique(norm[::-1], return_index=True)
            if last.size != idxs.size:
                keep = np.sort(idxs.size - 1 - last)
                idxs, ids = np.ascontiguousarray(idxs[keep]), np.ascontiguousarray(ids[keep])
        out = _lib.out_array(self.size, ids.dtype)
        self._ck(self._l.pfd_basins(self._h, _lib.ptr(idxs), idxs.size, _lib.dtype_code(idxs.dtype), _lib.ptr(ids),
                                    _lib.dtype_code(ids.dtype), _lib.ptr(out)))
        return out

    def strahler(self, mask=None):
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask)
            m = m.view(np.uint8) if m.dtype == np.bool_ else (m != 0).astype(np.uint8)
            if m.size != self.size:
                raise ValueError('"mask" size does not match.')
        out = _lib.out_array(self.size, np.uint8)
        self._ck(self._l.pfd_strahler(self._h, _lib.ptr(m), _lib.ptr(out)))
        return out

    def hand(self, drain, elevtn):
        d = np.ascontiguousarray(drain)
        d = d.view(np.uint8) if d.dtype == np.bool_ else (d == 1).astype(np.uint8)
        e = np.ascontiguousarray(elevtn)
        if e.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            e = e.astype(np.float64)  # integer DEMs: differences are exact in float64
        if d.size != self.size or e.size != self.size:
            raise ValueError('"elevtn" size does not match.')
        out = _lib.out_array(self.size, np.float64)
        self._ck(self._l.pfd_hand(self._h, _lib.ptr(d), _lib.ptr(e), _lib.dtype_code(e.dtype), _lib.ptr(out)))
        return out

    def fillnodata(self, data, nodata, direction="down", how="max"):
        data = np.ascontiguousarray(data)
        if data.size != self.size:
            raise ValueError('"data" size does not match.')
        if data.dtype == np.bool_:
            raise TypeError("fillnodata: boolean data is not supported")
        out = _lib.out_array(data.size, data.dtype)
        nd_f, nd_i, nd_is = nodata_args(nodata)
        self._ck(self._l.pfd_fillnodata(self._h, _lib.ptr(data), _lib.dtype_code(data.dtype), nd_f, nd_i, nd_is,
                                        0 if direction == "up" else 1, {"max": 0, "min": 1, "sum": 2}[how],
                                        _lib.ptr(out)))
        return out

    def main_upstream(self, uparea, upa_min=0.0, idx_dtype=np.int32):
        up = np.ascontiguousarray(uparea)
        if up.size != self.size:
            raise ValueError('"uparea" size does not match.')
        if up.dtype not in (np.dtype(np.int32), np.dtype(np.uint32), np.dtype(np.int64), np.dtype(np.float32),
                            np.dtype(np.float64)):
            up = up.astype(np.float64)
        idx_dtype = np.dtype(idx_dtype)
        fetch = np.dtype(np.int64) if idx_dtype == np.uint64 else idx_dtype
        out = _lib.out_array(self.size, fetch)
        self._ck(self._l.pfd_main_upstream(self._h, _lib.ptr(up), _lib.dtype_code(up.dtype), float(upa_min),
                                           _lib.ptr(out), _lib.dtype_code(fetch)))
        return out.astype(idx_dtype, copy=False)

    def upstream_matrix(self, idx_dtype=np.int32):
        """core.upstream_matrix -> (N, d) upstream indices in ascending order, padded with mv."""
        d = C.c_int64()
        self._ck(self._l.pfd_upstream_matrix(self._h, None, _lib.dtype_code(idx_dtype), 0, C.byref(d)))
        out = np.empty((self.size, d.value), dtype=idx_dtype)
        if d.value:
            self._ck(self._l.pfd_upstream_matrix(self._h, _lib.ptr(out), _lib.dtype_code(idx_dtype), d.value, C.byref(d)))
        return out

    def upstream_count(self, mask=None):
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask)
            m = m.view(np.uint8) if m.dtype == np.bool_ else (m != 0).astype(np.uint8)
            if m.size != self.size:
                raise ValueError('"mask" size does not match.')
        out = _lib.out_array(self.size, np.int8)
        self._ck(self._l.pfd_upstream_count(self._h, _lib.ptr(m), _lib.ptr(out)))
        return out

    def stream_order_classic(self, idxs_us_main, mask=None):
        um = np.ascontiguousarray(idxs_us_main)
        if um.dtype == np.uint64:
            um = um.astype(np.int64)
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask)
            m = m.view(np.uint8) if m.dtype == np.bool_ else (m != 0).astype(np.uint8)
        out = _lib.out_array(self.size, np.uint8)
        self._ck(self._l.pfd_stream_order_classic(self._h, _lib.ptr(um), _lib.dtype_code(um.dtype), _lib.ptr(m),
                                                  _lib.ptr(out)))
        return out

    def stream_distance(self, mask=None, hop_table=None):
        """hop_table None -> int32 cell counts, else float32 with the given [nrow, 3, 2] float32 hop lengths."""
        m = None
        if mask is not None:
            m = np.ascontiguousarray(mask)
            m = m.view(np.uint8) if m.dtype == np.bool_ else (m != 0).astype(np.uint8)
            if m.size != self.size:
                raise ValueError('"mask" size does not match.')
        real = hop_table is not None
        out = _lib.out_array(self.size, np.float32 if real else np.int32)
        tab = np.ascontiguousarray(hop_table, dtype=np.float32) if real else None
        self._ck(self._l.pfd_stream_distance(self._h, _lib.ptr(m), 1 if real else 0, _lib.ptr(tab), _lib.ptr(out)))
        return out

    def floodplains(self, drainh_init, elevtn):
        dh = np.ascontiguousarray(drainh_init, dtype=np.float32)
        e = np.ascontiguousarray(elevtn)
        if e.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
            e = e.astype(np.float64)
        if dh.size != self.size or e.size != self.size:
            raise ValueError('"elevtn" size does not match.')
        out = _lib.out_array(self.size, np.int8)
        self._ck(self._l.pfd_floodplains(self._h, _lib.ptr(dh), _lib.ptr(e), _lib.dtype_code(e.dtype), _lib.ptr(out)))
        return out

    # -- instrumentation
    @property
    def launches(self):
        return int(self._l.pfd_launch_count(self._h))

    def stage_ms(self, stage):
        return float(self._l.pfd_last_stage_ms(self._h, stage))
