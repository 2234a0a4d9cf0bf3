"""`from_array` / `FlwdirRaster`: the raster object API of the D8 hot path, mirroring
/root/reference/pyflwdir/pyflwdir.py (from_array :130-205, _get_idxs_dtype :105-127, FlwdirRaster.__init__
:211-273, idxs_seq :292-297, set_transform :318-337, to_array :341-360, basins :564-599, upstream_area :770-801,
hand :1485-1511, _check_data :1548-1559) with every kernel running on the GPU.

The D8 and PCRaster LDD flow-direction types share one parse kernel (different code table); CaMa-Flood "nextxy"
rasters are parsed by an element-wise kernel as long as every link stays inside the 8-neighbourhood.
"""
import pickle

import numpy as np

from . import _device, _lib, core_nextxy
from . import gis_utils as gis
from .flwdir import Flwdir, _not_in_scope
from .gis_utils import Affine

__all__ = ["FlwdirRaster", "from_array", "from_dem"]

FTYPES = ("d8", "ldd", "nextxy")  # pyflwdir.py:26-30
_MV = {"d8": np.uint8(247), "ldd": np.uint8(255)}  # core_d8.py:17, core_ldd.py:15


def _get_idxs_dtype(n):
    """Smallest integer dtype that can represent ``n`` indices (pyflwdir.py:105-127)."""
    if n < 2147483647:  # 2**31 - 1
        return np.int32
    elif n < 4294967294:  # 2**32 - 2
        return np.uint32
    return np.int64


def _is_d8_candidate(data):
    return isinstance(data, np.ndarray) and data.dtype == np.uint8 and data.ndim == 2


def from_array(data, ftype="infer", check_ftype=True, mask=None, transform=gis.IDENTITY, latlon=False,
               device=0, devices=None, **kwargs):
    """Flow direction raster array parsed to actionable format (GPU resident).

    Same signature and errors as the reference's `pyflwdir.from_array`; `device` (CUDA ordinal) and `devices` are the
    only extensions. The D8 raster is parsed by one CUDA kernel into the device flow graph; `idxs_ds`, `idxs_pit`,
    `idxs_seq`, `rank` are materialised on the host only when read.
    devices=[0, 1, ...] (D8 rasters): the raster is row-tiled over those GPUs (multigpu.MultiDeviceGraph): idxs_ds, rank,
    upstream_area(), basins(), stream_order(), accuflux(), hand() come from all of them over NCCL, bit-identical to one GPU.
    """
    infer = ftype == "infer"
    is_xy = core_nextxy.isformat(data)
    if infer:
        # the reference tries d8, ldd, nextxy in that order (pyflwdir.py:39-48)
        if not (_is_d8_candidate(data) or is_xy):
            raise ValueError("The flow direction type could not be inferred.")
        check_ftype = False
        candidates = ["nextxy"] if is_xy else ["d8", "ldd"]
    else:
        candidates = [ftype]
    if ftype == "nextxy" or (infer and is_xy):
        shape, ndim = data[0].shape, data[0].ndim
    else:
        ndim, shape = data.ndim, data.shape
    if ndim != 2:
        raise ValueError("The FlwdirRaster should be 2 dimensional")
    if not infer and ftype not in FTYPES:
        ftypes_str = '" ,"'.join(FTYPES)
        raise ValueError(f'Unknown flow direction type: "{ftype}", select from {ftypes_str}')
    invalid_msg = f'The flow direction data with type "{ftype}" is invalid.'
    dtype = _get_idxs_dtype(shape[0] * shape[1])
    dev = None
    if candidates == ["nextxy"]:
        if not is_xy:
            if check_ftype:
                raise ValueError(invalid_msg)
            raise TypeError("NEXTXY flwdir data not understood")
        nextx, nexty = data
        if infer or check_ftype:  # core_nextxy.isvalid (core_nextxy.py:86-103): the value tests run in the parse kernel
            ok = (isinstance(nextx, np.ndarray) and isinstance(nexty, np.ndarray) and nextx.dtype == "int32"
                  and nexty.dtype == "int32" and nextx.shape == nexty.shape)
            if not ok:
                raise ValueError("The flow direction type could not be inferred." if infer else invalid_msg)
        if mask is not None:
            if mask.shape != np.shape(data):  # the reference compares with the [2, nrow, ncol] stack (pyflwdir.py:185-188)
                raise ValueError('"mask" shape does not match with data shape')
            nextx, nexty = np.where(mask != 0, np.asarray(data), core_nextxy._mv)
        dev = _device.DeviceGraph(device)
        try:
            dev.parse_nextxy(nextx, nexty, check=infer or check_ftype)
        except ValueError as err:
            if getattr(err, "status", None) != _lib.ERR_INVALID_D8:
                raise
            raise ValueError("The flow direction type could not be inferred." if infer else invalid_msg) from None
        ftype = "nextxy"
        candidates = []
    elif not _is_d8_candidate(data):
        if check_ftype:
            raise ValueError(invalid_msg)
        raise ValueError("flow direction data must be a 2-D uint8 array")
    elif mask is not None and mask.shape != data.shape:
        raise ValueError('"mask" shape does not match with data shape')

    if dev is None:
        if devices is not None and len(devices) > 1:
            if candidates not in (["d8"], ["d8", "ldd"]):
                raise ValueError('devices=[...] supports ftype "d8"')
            from . import multigpu

            dev = multigpu.MultiDeviceGraph(devices)
            candidates = ["d8"]
        else:
            dev = _device.DeviceGraph(device if not devices else devices[0])
    for k, ft in enumerate(candidates):
        try:
            # illegal codes are refused even with check_ftype=False (the reference would mis-parse them)
            dev.parse_d8(data if mask is None else np.where(mask != 0, data, _MV[ft]), ftype=ft)
            ftype = ft
            break
        except ValueError as err:
            if getattr(err, "status", None) != _lib.ERR_INVALID_D8:
                raise
            if k + 1 == len(candidates):
                if infer:
                    raise ValueError("The flow direction type could not be inferred.") from None
                raise ValueError(invalid_msg) from None
    idxs_pit = dev.fetch(_lib.ARR_PITS, dtype)
    is_outlet = dev.fetch(_lib.ARR_PIT_IS_OUTLET)
    idxs_outlet = idxs_pit[is_outlet != 0]  # pits whose code is 0/255 (pyflwdir.py:193)
    return FlwdirRaster(
        idxs_ds=None, idxs_pit=idxs_pit, idxs_outlet=idxs_outlet, shape=shape, ftype=ftype, transform=transform,
        latlon=latlon, _dev=dev, _idx_dtype=dtype, **kwargs,
    )


def from_dem(data, nodata=-9999.0, max_depth=-1.0, transform=gis.IDENTITY, latlon=False, outlets="edge", **kwargs):
    """Flow direction raster derived from digital elevation data (pyflwdir.py:51-102): dem.fill_depressions on the GPU
    (outlets at the edge of the valid cells, or only the lowest one with outlets="min"; depressions filled to their pour
    point) followed by from_array on the resulting D8 raster. `max_depth >= 0` is not implemented (see dem.fill_depressions)."""
    from . import dem

    d8 = dem.fill_depressions(data, nodata=nodata, max_depth=max_depth, outlets=outlets, device=kwargs.get("device", 0))[1]
    return from_array(d8, ftype="d8", check_ftype=False, transform=transform, latlon=latlon, **kwargs)


class FlwdirRaster(Flwdir):
    """Flow direction raster array parsed to general actionable format."""

    def __init__(self, idxs_ds, shape, ftype, idxs_pit=None, idxs_outlet=None, idxs_seq=None, nnodes=None,
                 transform=gis.IDENTITY, latlon=False, cache=True, device=0, _dev=None, _idx_dtype=None):
        shape = tuple(int(s) for s in shape)
        if _dev is None:
            idxs_ds = np.asarray(idxs_ds)
            size = idxs_ds.size
        else:
            size = _dev.size
        super().__init__(idxs_ds=idxs_ds, idxs_pit=idxs_pit, idxs_outlet=idxs_outlet, idxs_seq=idxs_seq,
                         nnodes=nnodes, cache=cache, _dev=_dev, _idx_dtype=_idx_dtype, _size=size)
        if ftype not in FTYPES:
            ftypes_str = '" ,"'.join(FTYPES)
            raise ValueError(f'Unknown flow direction type: "{ftype}", select from {ftypes_str}')
        self.ftype = ftype
        if len(shape) != 2 or shape[0] * shape[1] != self.size:
            raise ValueError(f"Invalid FlwdirRaster: shape {shape} does not match size {self.size}")
        self.shape = shape
        self.set_transform(transform, latlon)
        if self._dev is None:  # public constructor path: upload the index array
            self._dev = _device.DeviceGraph(device)
            self._dev.load_idxs_ds(self._idxs_ds, shape)
            if self._pit is None:
                self._pit = self._dev.fetch(_lib.ARR_PITS, self._idx_dtype)
        # check validity (flwdir.py:125-127)
        if self.idxs_pit.size == 0:
            raise ValueError("Invalid FlwdirRaster: no pits found")

    def _raster_shape(self):
        return self.shape

    @property
    def _dict(self):
        return {
            "ftype": self.ftype,
            "shape": self.shape,
            "nnodes": self.nnodes,
            "transform": self.transform,
            "latlon": self.latlon,
            "idxs_ds": self.idxs_ds,
            "idxs_seq": self._seq,
            "idxs_pit": self._pit,
        }

    @property
    def ncells(self):
        return self.nnodes

    @property
    def idxs_seq(self):
        """Linear indices of valid cells ordered from down- to upstream (pyflwdir.py:292-297: "walk", except for
        nextxy rasters, which the reference orders with np.argsort of the rank; ties inside a rank level then follow
        numpy's unstable sort. The device sweeps always run over the "walk" sequence.)"""
        if self._seq is None:
            self.order_cells(method="walk" if self.ftype != "nextxy" else "sort")
        return self._seq

    @property
    def area(self):
        """Cell area [m2]"""
        if "area" in self._cached:
            area = self._cached["area"]
        else:
            area = gis.area_grid(self.transform, self.shape, self.latlon, unit="m2")
            if self.cache:
                self._cached.update(area=area)
        return area

    @property
    def bounds(self):
        """Returns the raster bounding box [xmin, ymin, xmax, ymax] (pyflwdir.py:409-412 -> gis_utils.array_bounds)."""
        w, n = self.transform.xoff, self.transform.yoff
        e, s = self.transform * (self.shape[1], self.shape[0])
        return np.array((w, s, e, n), dtype=np.float64)

    @property
    def extent(self):
        """Returns the raster extent in cartopy format [xmin, xmax, ymin, ymax]."""
        xmin, ymin, xmax, ymax = self.bounds
        return np.array([xmin, xmax, ymin, ymax], dtype=np.float64)

    @property
    def distnc(self):
        """Distance to outlet [m] (pyflwdir.py:420-429)"""
        if "distnc" in self._cached:
            distnc = self._cached["distnc"]
        else:
            distnc = self.stream_distance(unit="m")
            if self.cache:
                self._cached.update(distnc=distnc)
        return distnc

    # ------------------------------------------------------------------ set / modify
    def set_transform(self, transform, latlon=False):
        """Set transform affine (pyflwdir.py:318-337)."""
        if not isinstance(transform, Affine):
            try:
                transform = Affine(*tuple(transform)[:6])
            except TypeError:
                raise ValueError("Invalid transform.")
        self.transform = transform
        self.latlon = latlon

    def add_pits(self, idxs=None, xy=None, streams=None):
        idxs1 = self._check_idxs_xy(idxs, xy, streams)
        super().add_pits(idxs=idxs1)

    # ------------------------------------------------------------------ export
    def to_array(self, ftype=None):
        """Return 2D flow direction raster (core_d8.to_array, core_d8.py:86-102)."""
        if ftype is None:
            ftype = self.ftype
        if ftype == "d8":
            return self._dev.fetch(_lib.ARR_D8).reshape(self.shape)
        if ftype == "ldd":
            return self._dev.fetch(_lib.ARR_LDD).reshape(self.shape)
        if ftype == "nextxy":
            return self._dev.fetch(_lib.ARR_NEXTXY).reshape((2,) + self.shape)
        raise ValueError(f'ftype "{ftype}" unknown')

    @staticmethod
    def load(fn):
        """Load serialized FlwdirRaster object from file"""
        with open(fn, "rb") as handle:
            kwargs = pickle.load(handle)
        return FlwdirRaster(**kwargs)

    # ------------------------------------------------------------------ spatial helpers
    def index(self, xs, ys, **kwargs):
        """Linear cell indices of x, y coordinates (pyflwdir.py:388-406 -> gis_utils.coords_to_idxs,
        gis_utils.py:304-338): raises IndexError for coordinates outside the raster, like the reference."""
        xs, ys = np.atleast_1d(xs), np.atleast_1d(ys)
        cols, rows = ~self.transform * (xs, ys)
        cols, rows = np.floor(cols).astype(np.int64), np.floor(rows).astype(np.int64)
        nrow, ncol = self.shape
        if np.any((rows < 0) | (rows >= nrow) | (cols < 0) | (cols >= ncol)):
            raise IndexError("XY coordinates outside domain")
        return rows * ncol + cols

    def xy(self, idxs, **kwargs):
        """Cell-centre x, y coordinates of linear indices (pyflwdir.py:408-424 -> gis_utils.idxs_to_coords,
        gis_utils.py:264-301): raises IndexError for indices outside the raster, like the reference."""
        idxs = np.atleast_1d(idxs)
        nrow, ncol = self.shape
        if np.any((idxs < 0) | (idxs >= nrow * ncol)):
            raise IndexError("idxs coordinates outside domain")
        r, c = idxs // ncol, idxs % ncol
        return self.transform * (c + 0.5, r + 0.5)

    # ------------------------------------------------------------------ local traces
    def _trace_args(self, unit, direction):
        unit = str(unit).lower()
        if unit not in ["m", "cell"]:
            raise ValueError(f'Unknown unit: {unit}, select from ["m", "cell"].')
        direction = str(direction).lower()
        if direction not in ["up", "down"]:
            msg = 'Unknown flow direction: {direction}, select from ["up", "down"].'
            raise ValueError(msg)
        hop = gis.hop_length_table(self.shape[0], self.transform, self.latlon, dtype=np.float64) if unit == "m" else None
        return direction, hop

    def path(self, idxs=None, xy=None, mask=None, max_length=None, unit="cell", direction="down"):
        """Returns paths of indices in down- or upstream direction from the starting points until a pit / headwater, a
        True cell in mask (included) or max_length is exceeded (pyflwdir.py:443-500 -> core.path, core.py:400-438)."""
        direction, hop = self._trace_args(unit, direction)
        paths, _, dist = self._dev.trace(
            self._check_idxs_xy(idxs, xy), direction, self.idxs_us_main if direction == "up" else None,
            self._check_data(mask, "mask", optional=True), max_length, hop, paths_dtype=self._idx_dtype)
        return paths, dist

    def snap(self, idxs=None, xy=None, mask=None, max_length=None, unit="cell", direction="down"):
        """Returns the last index of the trace from every starting point and the distance to it
        (pyflwdir.py:502-560 -> core.snap, core.py:441-480: indices in the dtype of idxs, float32 distances)."""
        direction, hop = self._trace_args(unit, direction)
        idxs0 = self._check_idxs_xy(idxs, xy)
        _, ends, dist = self._dev.trace(
            idxs0, direction, self.idxs_us_main if direction == "up" else None,
            self._check_data(mask, "mask", optional=True), max_length, hop)
        return ends.astype(idxs0.dtype), dist.astype(np.float32)

    def inflow_idxs(self, region):
        """Returns linear indices of most upstream cells within region (pyflwdir.py:804-818 -> core.inflow_idxs)."""
        return self._dev.inflow_idxs(self._check_data(region, "region"), self._idx_dtype)

    def outflow_idxs(self, region):
        """Returns linear indices of most downstream cells within region (pyflwdir.py:820-835 -> core.outflow_idxs)."""
        return self._dev.outflow_idxs(self._check_data(region, "region"), self._idx_dtype)

    # ------------------------------------------------------------------ regions
    def basin_outlets(self, basins):
        """Returns the linear index of the outlet cell of `basins` (pyflwdir.py:720-740 -> regions.region_outlets)."""
        return self._dev.region_outlets(self._check_data(basins, "basins"), self._idx_dtype)

    def basin_bounds(self, basins=None, **kwargs):
        """Returns the basin labels, their bounding boxes [xmin, ymin, xmax, ymax] and the total bounding box
        (pyflwdir.py:694-718 -> regions.region_bounds, regions.py:89-129; the label extents come from the device)."""
        regions = self._check_data(basins, "basins", flatten=False, **kwargs)
        if regions.ndim != 2:
            raise ValueError('The "regions" array should be two dimensional')
        lbs, sl = self._dev.region_slices(regions)
        if lbs.size == 0:
            raise ValueError("No regions found in data")
        xres, yres = self.transform[0], self.transform[4]
        lons, lats = gis.affine_to_coords(self.transform, regions.shape)
        iy = np.array([0, -1])
        ix = iy.copy()
        if yres < 0:
            iy = iy[::-1]
        if xres < 0:
            ix = ix[::-1]
        dx = np.abs(xres) / 2
        dy = np.abs(yres) / 2
        # lons[xslice][ix], lats[yslice][iy] of regions.py:123-125 for all labels at once (first / last cell of the slice)
        r0, r1, c0, c1 = sl[:, 0], sl[:, 1] - 1, sl[:, 2], sl[:, 3] - 1
        xends, yends = (c0, c1), (r0, r1)
        xmin, xmax = lons[xends[ix[0]]], lons[xends[ix[1]]]
        ymin, ymax = lats[yends[iy[0]]], lats[yends[iy[1]]]
        bboxs = np.stack([xmin - dx, ymin - dy, xmax + dx, ymax + dy], axis=1)
        total_bbox = np.hstack([bboxs[:, :2].min(axis=0), bboxs[:, 2:].max(axis=0)])
        return lbs, bboxs, total_bbox

    def interbasin_mask(self, region, stream=None):
        """Returns most downstream contiguous area within region (pyflwdir.py:742-766 -> basins.interbasin_mask)."""
        mask = self._dev.interbasin_mask(self._check_data(region, "region"), self._check_data(stream, "stream", optional=True))
        return mask.reshape(self.shape)

    # ------------------------------------------------------------------ vectorize
    def vectorize(self, mask=None, xs=None, ys=None, direction="down", **kwargs):
        """Returns each flow direction as a linestring geo-feature (pyflwdir.py:865-892 -> core.flwdir_tuples, core.py:266-275:
        one [cell, next cell] pair per valid cell)."""
        nxt = self.idxs_ds if direction == "down" else self.idxs_us_main
        valid = nxt != self._mv
        m = self._check_data(mask, "mask", optional=True)
        if m is not None:
            valid &= m == 1
        idx0 = np.flatnonzero(valid).astype(nxt.dtype)
        pairs = np.stack([idx0, nxt[idx0]], axis=1)
        return self.geofeatures(list(pairs), xs=xs, ys=ys, **kwargs)

    def streams(self, mask=None, min_sto=1, xs=None, ys=None, idxs_out=None, max_len=0, direction="up", **kwargs):
        """Returns a list of stream segments (between two confluences) as linestring geo-features
        (pyflwdir.py:894-974 -> streams.streams, streams.py:131-188, on the device; gis_utils.features on the host)."""
        if mask is not None:
            mask = self._check_data(mask, "mask")
        elif min_sto > 1:
            strord = self._check_data(kwargs.get("strord"), "strord")
            mask = strord >= min_sto
            kwargs.update(strord=strord)  # add strord column
        if idxs_out is not None:
            raise NotImplementedError("streams(idxs_out=...) needs subgrid.segment_indices, which is outside the D8 hot path "
                                      "that pyflwdir_b200 accelerates")
        idxs = self._dev.streams(mask, max_len, self._idx_dtype)
        return self.geofeatures(idxs, xs=xs, ys=ys, **kwargs)

    def geofeatures(self, flowpaths, xs=None, ys=None, **kwargs):
        """Returns geo-features of flowpaths defined by a list of arrays of linear indices (pyflwdir.py:976-1011)."""
        return gis.features(flowpaths=flowpaths, xs=self._check_data(xs, "xs", optional=True),
                            ys=self._check_data(ys, "ys", optional=True), transform=self.transform, shape=self.shape, **kwargs)

    # ------------------------------------------------------------------ basins
    def basins(self, idxs=None, xy=None, ids=None, **kwargs):
        """(Sub)basin map with a unique ID for every (sub)basin (pyflwdir.py:564-599 -> basins.basins)."""
        if idxs is None and xy is None:  # full basins / includes edge-pits
            idxs_dev = None
            n_idxs = self.idxs_pit.size
        else:
            idxs_dev = self._check_idxs_xy(idxs, xy, **kwargs)
            n_idxs = idxs_dev.size
        if ids is not None:
            ids = np.atleast_1d(ids).ravel()
            if ids.size != n_idxs:
                raise ValueError("IDs size does not match size of idxs.")
            elif np.any(ids == 0):
                raise ValueError("IDs cannot contain a value zero.")
        if idxs_dev is None and ids is None:
            basids = self._dev.basins()
        else:
            if idxs_dev is None:
                idxs_dev = self.idxs_pit
            if ids is None:
                ids = np.arange(1, n_idxs + 1, dtype=np.uint32)  # basins.py:14-15
            basids = self._dev.basins(idxs_dev, ids)
        return basids.reshape(self.shape)

    # ------------------------------------------------------------------ accumulate
    def subbasins_streamorder(self, strord=None, mask=None, min_sto=-2):
        """Returns map with basin IDs, with one basin for each stream with a minimal stream order (pyflwdir.py:601-629 ->
        basins.subbasins_streamorder): (int32 map, linear indices of the subbasin outlets)."""
        subbas, idxs_out = self._dev.subbasins_streamorder(
            self._check_data(strord, "strord"), self._check_data(mask, "mask", optional=True), min_sto,
            idx_dtype=self._idx_dtype)
        return subbas.reshape(self.shape), idxs_out

    def subbasins_pfafstetter(self, depth=1, uparea=None, upa_min=0.0):
        """Returns the pfafstetter subbasins and the linear indices of their outlet cells
        (pyflwdir.py:631-663 -> basins.subbasins_pfafstetter, basins.py:106-191)."""
        uparea = self._check_data(uparea, "uparea")
        mask = uparea >= upa_min if upa_min is not None else None
        subbas, idxs_out = self._dev.subbasins_pfafstetter(self.idxs_us_main, uparea, mask=mask, depth=depth,
                                                           idx_dtype=self._idx_dtype)
        return subbas.reshape(self.shape), idxs_out

    def subbasins_area(self, area_min, uparea=None):
        """Returns map with basin IDs, with a minimal area of `area_min` (pyflwdir.py:665-692 -> basins.subbasins_area):
        (uint32 map, linear indices of the subbasin outlets)."""
        subbas, idxs_out = self._dev.subbasins_area(
            self.idxs_us_main, self._check_data(uparea, "uparea", unit="km2"), area_min, idx_dtype=self._idx_dtype)
        return subbas.reshape(self.shape), idxs_out

    def upstream_area(self, unit="cell"):
        """Upstream area map (pyflwdir.py:770-801). "cell": int32 counts by the dedicated device sweep; other
        units accumulate the cell-area grid in its own dtype (float64 if latlon, else float32)."""
        unit = str(unit).lower()
        if unit not in gis.AREA_FACTORS:
            fstr = '", "'.join(gis.AREA_FACTORS.keys())
            raise ValueError(f'Unknown unit: {unit}, select from "{fstr}".')
        if unit == "cell":
            return self._dev.upstream_area_cells().reshape(self.shape)  # -9999 fill is fused in the kernel
        area = self.area.ravel() / gis.AREA_FACTORS[unit]
        uparea = self._dev.accuflux(area, -9999, "up")
        uparea[~self.mask.ravel()] = -9999
        return uparea.reshape(self.shape)

    # ------------------------------------------------------------------ streams
    def stream_distance(self, mask=None, unit="cell"):
        """Distance to the outlet or to the next downstream True cell of `mask` (pyflwdir.py:837-863 ->
        streams.stream_distance): int32 cell counts for unit="cell", float32 metres for unit="m"."""
        unit = str(unit).lower()
        if unit not in ["m", "cell"]:
            raise ValueError(f'Unknown unit: {unit}, select from "m", "cell"')
        mask = self._check_data(mask, "mask", optional=True)
        table = gis.hop_length_table(self.shape[0], self.transform, self.latlon) if unit != "cell" else None
        return self._dev.stream_distance(mask, table).reshape(self.shape)

    # ------------------------------------------------------------------ elevation
    def hand(self, drain, elevtn):
        """Height above the nearest drain (pyflwdir.py:1485-1511 -> dem.height_above_nearest_drain)."""
        hand = self._dev.hand(self._check_data(drain, "drain"), self._check_data(elevtn, "elevtn"))
        return hand.reshape(self.shape)

    def floodplains(self, elevtn, uparea=None, upa_min=1000, b=0.3):
        """Floodplain boundaries from a HAND threshold that scales with upstream area, h ~ A**b
        (pyflwdir.py:1513-1546 -> dem.floodplains, dem.py:333-379). int8: 1 floodplain, 0 not, -1 outside the sequence."""
        elev = self._check_data(elevtn, "elevtn")
        upa = self._check_data(uparea, "uparea", unit="km2")
        # uparea ** b for the drain cells only, with numpy scalars exactly as the reference's Python loop evaluates it
        drain = np.flatnonzero(upa >= upa_min)
        drainh = np.full(self.size, -9999.0, dtype=np.float32)
        if drain.size:
            drainh[drain] = np.array([u ** b for u in upa[drain]], dtype=np.float64).astype(np.float32) \
                if upa.dtype != np.float32 else np.array([u ** b for u in upa[drain]], dtype=np.float32)
        return self._dev.floodplains(drainh, elev).reshape(self.shape)

    # ------------------------------------------------------------------ shortcuts
    def _check_data(self, data, name, optional=False, flatten=True, **kwargs):
        if data is None and optional:
            return
        if data is None:
            if name == "uparea":
                data = self.upstream_area(**kwargs)
            elif name == "basins":
                data = self.basins(**kwargs)
            elif name == "strord":
                data = self.stream_order(**kwargs)
        return super()._check_data(data, name, optional, flatten=flatten)

    def _check_idxs_xy(self, idxs=None, xy=None, streams=None):
        if (xy is not None and idxs is not None) or (xy is None and idxs is None):
            raise ValueError("Either idxs or xy should be provided.")
        elif xy is not None:
            idxs = self.index(*xy)
        return super()._check_idxs_xy(idxs, streams)

    for _name in ("dem_dig_d4", "upscale", "upscale_error", "ucat_outlets", "ucat_area", "ucat_volume",
                  "subgrid_rivlen", "subgrid_rivslp", "subgrid_rivavg", "subgrid_rivmed"):
        locals()[_name] = _not_in_scope(_name)
    del _name
