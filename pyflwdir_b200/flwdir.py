"""`Flwdir`: the format-agnostic part of the object API, mirroring /root/reference/pyflwdir/flwdir.py for the
hot path only (properties :148-227, order_cells :231-250, add_pits/repair_loops :260-286, dump/load :290-306,
stream_order :508-547, accuflux :567-602, _check_data :782-803).

All heavy lifting happens on the GPU through `_device.DeviceGraph`; host copies of idxs_ds / idxs_pit / idxs_seq /
rank are materialised lazily, the first time a property is read. Vector networks (`from_dataframe`) and the
methods outside SURVEY.md §8 are not provided here.
"""
import pickle
import pprint

import numpy as np

from . import _lib

__all__ = ["Flwdir"]

_mv = np.intp(-1)  # core._mv, pyflwdir/core.py:12


def _not_in_scope(name):
    def method(self, *args, **kwargs):
        raise NotImplementedError(
            f"{type(self).__name__}.{name} is outside the D8 hot path that pyflwdir_b200 accelerates "
            "(SURVEY.md §8); use Deltares/pyflwdir for it -- idxs_ds / idxs_seq / idxs_pit from this object are "
            "bit-identical to the reference's and can be handed to it."
        )

    method.__name__ = name
    return method


class Flwdir(object):
    """Flow direction parsed to general actionable format (device-resident)."""

    def __init__(self, idxs_ds, idxs_pit=None, idxs_outlet=None, idxs_seq=None, nnodes=None, cache=True,
                 _dev=None, _idx_dtype=None, _size=None):
        # dimension
        self.size = int(_size if _size is not None else idxs_ds.size)
        if self.size <= 1:
            raise ValueError(f"Invalid FlwdirRaster: size {self.size}")
        self.shape = self.size
        # data (host copies are lazy when the graph was parsed on the device)
        self._idxs_ds = idxs_ds
        self._idx_dtype = np.dtype(_idx_dtype if _idx_dtype is not None else idxs_ds.dtype)
        self._pit = idxs_pit
        self.idxs_outlet = idxs_outlet
        self._seq = idxs_seq
        self._nnodes = nnodes
        self._dev = _dev
        # either -1 for int, 4294967295 for uint32, or 18446744073709551615 for uint64 (flwdir.py:111-117)
        self._mv = _mv
        if self._idx_dtype == np.uint32:
            self._mv = np.uint32(_mv)
        if self._idx_dtype == np.uint64:
            self._mv = np.uint64(_mv)
        self.cache = cache
        self._cached = dict()

    # ------------------------------------------------------------------ representation
    def __str__(self):
        return pprint.pformat(self._dict)

    def __getitem__(self, idx):
        return self.idxs_ds[idx]

    @property
    def _dict(self):
        return {"nnodes": self.nnodes, "idxs_ds": self.idxs_ds, "idxs_seq": self._seq, "idxs_pit": self._pit}

    # ------------------------------------------------------------------ properties
    @property
    def idxs_ds(self):
        """Linear indices of downstream cell."""
        if self._idxs_ds is None:
            self._idxs_ds = self._dev.fetch(_lib.ARR_IDXS_DS, self._idx_dtype)
        return self._idxs_ds

    @property
    def idxs_us_main(self):
        """Linear indices of main upstream cell, i.e. the upstream cell with the largest contributing area."""
        if "idxs_us_main" in self._cached:
            return self._cached["idxs_us_main"]
        return self.main_upstream()

    @property
    def idxs_seq(self):
        """Linear indices of valid cells ordered from down- to upstream."""
        if self._seq is None:
            self.order_cells(method="walk")
        return self._seq

    @property
    def idxs_pit(self):
        """Linear indices of pits/outlets."""
        if self._pit is None:
            self._pit = self._dev.fetch(_lib.ARR_PITS, self._idx_dtype)
        return self._pit

    @property
    def nnodes(self):
        """Number of valid cells."""
        if self._nnodes is None:
            self._nnodes = int(self._dev.order()[0])
        return self._nnodes

    @property
    def rank(self):
        """Cell Rank, i.e. distance to the outlet in no. of cells."""
        if "rank" in self._cached:
            rank = self._cached["rank"]
        else:
            rank = self._dev.fetch(_lib.ARR_RANK).reshape(self.shape)
            if self.cache:
                self._cached.update(rank=rank)
        return rank

    @property
    def isvalid(self):
        """True if the flow direction map is valid."""
        self._cached.pop("rank", None)
        return np.all(self.rank != -1)

    @property
    def mask(self):
        """Boolean array of valid cells in flow direction raster."""
        if self._idxs_ds is not None:
            return self._idxs_ds != self._mv
        return self._dev.fetch(_lib.ARR_D8) != np.uint8(247)

    @property
    def area(self):
        """Cell area [m]"""
        if "area" in self._cached:
            return self._cached["area"]
        return np.ones(self.size, dtype=np.float32)

    @property
    def n_upstream(self):
        """Number of upstream connection"""
        return self._dev.fetch(_lib.ARR_N_UPSTREAM).reshape(self.shape)

    # ------------------------------------------------------------------ set / modify
    def order_cells(self, method="sort"):
        """Order cells from down- to upstream ("walk": the device BFS that reproduces core.idxs_seq exactly;
        "sort": argsort of the device rank, flwdir.py:241-244)."""
        if method == "sort":
            rnk = self._dev.fetch(_lib.ARR_RANK)
            n = int(np.sum(rnk >= 0))
            self._seq = np.argsort(rnk)[-n:].astype(self._idx_dtype) if n else np.empty(0, self._idx_dtype)
        elif method == "walk":
            self._seq = self._dev.fetch(_lib.ARR_SEQ, self._idx_dtype)
        else:
            raise ValueError(f'Invalid method {method}, select from ["walk", "sort"]')
        self._nnodes = self._seq.size

    def _reload_device(self):
        """Push a (mutated) host idxs_ds back to the device."""
        self._dev.load_idxs_ds(self._idxs_ds, self._raster_shape())
        self._cached.clear()

    def _raster_shape(self):
        raise NotImplementedError

    def main_upstream(self, uparea=None):
        """Index of the upstream cell with the largest upstream area, mv at headwaters (flwdir.py:252-258)."""
        idxs_us_main = self._dev.main_upstream(self._check_data(uparea, "uparea"), 0.0, self._idx_dtype)
        if self.cache:
            self._cached.update(idxs_us_main=idxs_us_main)
        return idxs_us_main

    def add_pits(self, idxs=None, streams=None):
        """Add pits to the flow direction (flwdir.py:260-279)."""
        idxs1 = self._check_idxs_xy(idxs, streams=streams)
        ids = self.idxs_ds  # materialise
        pits = self.idxs_pit
        ids[idxs1] = idxs1
        self._pit = np.unique(np.concatenate([pits, np.asarray(idxs1, dtype=pits.dtype)]))
        self._seq = None
        self._nnodes = None
        self._reload_device()

    def repair_loops(self):
        """Repair loops by setting a pit at every cell which does not drain to a pit (flwdir.py:281-286)."""
        repair_idx = np.flatnonzero(self.rank.ravel() == -1).astype(self._idx_dtype)
        if repair_idx.size > 0:
            self.add_pits(repair_idx)

    # ------------------------------------------------------------------ IO
    def dump(self, fn):
        """Serialize object to file using pickle library."""
        with open(fn, "wb") as handle:
            pickle.dump(self._dict, handle, protocol=-1)

    # ------------------------------------------------------------------ sweeps
    def stream_order(self, type="strahler", mask=None):
        """Strahler stream order map (streams.strahler_order); result cached under "strord" irrespective of
        `mask` when cache=True, as the reference does (flwdir.py:537-542)."""
        mask = self._check_data(mask, "mask", optional=True)
        if type.lower() == "strahler":
            if "strord" in self._cached:
                strord = self._cached["strord"]
            else:
                strord = self._dev.strahler(mask)
                if self.cache:
                    self._cached.update(strord=strord)
        elif type.lower() == "classic":
            strord = self._dev.stream_order_classic(self.idxs_us_main, mask)
        else:
            raise ValueError(f'Unknown stream order type: "{type}"')
        return strord.reshape(self.shape)

    def upstream_sum(self, data, mv=-9999):
        """Returns sum of next upstream values (flwdir.py:412-433 -> arithmetics.upstream_sum)."""
        dflat = self._check_data(data, "data")
        return self._dev.upstream_sum(dflat, mv).reshape(np.shape(data))

    def moving_average(self, data, n, weights=None, restrict_strord=False, strord=None, nodata=-9999.0):
        """Take the moving weighted average over the flow direction network (flwdir.py:435-470 ->
        arithmetics.moving_average): n up- and n downstream neighbours along the main stem."""
        dout = self._dev.moving_average(
            self._check_data(data, "data"), self._check_data(weights, "weights", optional=True), n, self.idxs_us_main,
            strord=self._check_data(strord, "strord", optional=not restrict_strord), nodata=nodata)
        return dout.reshape(np.shape(data))

    def moving_median(self, data, n, restrict_strord=False, strord=None, nodata=-9999.0):
        """Take the moving median over the flow direction network (flwdir.py:472-505 -> arithmetics.moving_median)."""
        dout = self._dev.moving_median(
            self._check_data(data, "data"), n, self.idxs_us_main,
            strord=self._check_data(strord, "strord", optional=not restrict_strord), nodata=nodata)
        return dout.reshape(np.shape(data))

    # ------------------------------------------------------------------ rivers
    @property
    def distnc(self):
        """Distance to outlet [m] (flwdir.py:206-213)"""
        if "distnc" in self._cached:
            return self._cached["distnc"]
        return np.ones(self.size, dtype=np.float32)

    def classify_estuaries(self, elevtn, rivwth, rivdst=None, min_convergence=1e-2, max_elevtn=0):
        """Classifies estuaries based on a minimum width convergence (flwdir.py:666-696 -> rivers.classify_estuary,
        rivers.py:11-53): int8, >= 1 where estuary, 2 at the upstream end of an estuary. Flat, like the reference."""
        rivdst = self.distnc if rivdst is None else rivdst
        rivdst = self._check_data(rivdst, "rivdst")
        rivwth = self._check_data(rivwth, "rivwth")
        elevtn = self._check_data(elevtn, "elevtn")
        est = np.zeros(self.size, np.int8)
        pits = self.idxs_pit
        est[pits[elevtn[pits] <= max_elevtn]] = 1  # rivers.py:38-39
        return self._dev.classify_estuary(est, rivdst, rivwth, min_convergence)

    def river_depth(self, *args, **kwargs):
        """Out of scope (rivers / hydraulics): element-wise host numpy in the reference
        (flwdir.py:698-778). Hand `idxs_ds` / `idxs_seq` of this object to the unmodified reference for it."""
        raise NotImplementedError("river_depth is outside the D8 hot path that pyflwdir_b200 accelerates; use the reference "
                                  "on this object's idxs_ds / idxs_seq (they are bit-identical to the reference's)")

    # ------------------------------------------------------------------ local methods
    def path(self, idxs=None, mask=None, max_length=None, direction="down"):
        """Returns paths of indices in down- or upstream direction from the starting points until a pit / headwater,
        a True cell in mask (included) or max_length [cells] is reached (flwdir.py:309-356 -> core.path)."""
        direction = str(direction).lower()
        if direction not in ["up", "down"]:
            msg = 'Unknown flow direction: {direction}, select from ["up", "down"].'
            raise ValueError(msg)
        paths, _, dist = self._dev.trace(
            np.atleast_1d(idxs).ravel(), direction, self.idxs_us_main if direction == "up" else None,
            self._check_data(mask, "mask", optional=True), max_length, None, paths_dtype=self._idx_dtype)
        return paths, dist

    def downstream(self, data):
        """Returns next downstream value (flwdir.py:394-410)."""
        dflat = self._check_data(data, "data")
        return self._dev.downstream(dflat).reshape(np.shape(data))

    def fillnodata(self, data, nodata, direction="down", how="max"):
        """Returns data where cells with nodata value have been filled with the nearest up- or downstream valid
        neighbor value (flwdir.py:360-392)."""
        direction = str(direction).lower()
        dflat = self._check_data(data, "data")
        if direction not in ("up", "down"):
            raise ValueError('Unknown flow direction: {direction}, select from ["up", "down"].')
        if how not in ("min", "max", "sum"):
            raise ValueError(f'Unknown how: "{how}", select from ["min", "max", "sum"].')
        dout = self._dev.fillnodata(dflat, nodata, direction, how)
        return dout.reshape(np.shape(data) if np.size(data) == self.size else self.shape)

    def upstream_area(self):
        """Upstream area from the cached/unit cell area (flwdir.py:549-565)."""
        uparea = self._dev.accuflux(self.area.ravel(), -9999, "up")
        uparea[~self.mask.ravel()] = -9999
        return uparea.reshape(self.shape)

    def accuflux(self, data, nodata=-9999, direction="up"):
        """Accumulated data values along the flow directions (flwdir.py:567-602)."""
        if direction not in ("up", "down"):
            raise ValueError('Unknown flow direction: {direction}, select from ["up", "down"].')
        accu = self._dev.accuflux(self._check_data(data, "data"), nodata, direction)
        return accu.reshape(np.shape(data) if np.size(data) == self.size else self.shape)

    # ------------------------------------------------------------------ shortcuts
    def _check_data(self, data, name, optional=False, flatten=True, **kwargs):
        """check data shape and size; by default return flattened array (flwdir.py:782-803)"""
        if data is None and optional:
            return
        if data is None:
            if name == "uparea":
                data = self.upstream_area(**kwargs)
            elif name == "strord":
                data = self.stream_order(**kwargs)
        data = np.atleast_1d(data)
        if flatten:
            if data.size == 1:
                data = np.full(self.size, data, dtype=data.dtype)
            elif data.size != self.size:
                raise ValueError(f'"{name}" size does not match.')
            return data.ravel()
        else:
            if data.size == 1:
                data = np.full(self.shape, data, dtype=data.dtype)
            elif data.shape != self.shape:
                raise ValueError(f'"{name}" shape does not match.')
            return data

    def _check_idxs_xy(self, idxs, streams=None):
        idxs = np.atleast_1d(idxs).ravel()
        # snap to streams (flwdir.py:805-811)
        streams = self._check_data(streams, "streams", optional=True)
        if streams is not None:
            idxs = self.snap(idxs=idxs, mask=streams)[0]
        return idxs

    # ------------------------------------------------------------------ not in scope
    for _name in ("smooth_rivlen", "dem_adjust"):
        locals()[_name] = _not_in_scope(_name)
    del _name
