"""`from_array(..., devices=[0, 1, ...])`: the drop-in surface over SEVERAL GPUs of one node, in one process.

The raster is cut into row blocks, one per device (`tiled.split_rows`); every block lives on its own GPU and is driven by
its own host thread (ctypes releases the GIL inside libpfd_b200), the ranks meeting over NCCL exactly as the
one-process-per-GPU path does (`tiled.RowBlockSolver`):

  * parse + rank + upstream_area("cell") + basins()   pfd_d8_flow_all_tiled   (one all-gather + one all-reduce)
  * stream_order() (unmasked), accuflux(direction="up"), upstream_area(unit != "cell"), hand()
                                                      pfd_sweep_tiled         (halo rounds, ncclSend / ncclRecv)

Everything else of the API (ordering, custom outlets, masks, traces, ...) runs on devices[0], where the whole raster is
parsed as well. Results are bit-identical to the single-GPU path (BASELINE.json config 4).
"""
import threading

import numpy as np

from . import _device, _lib, tiled


class MultiDeviceGraph(_device.DeviceGraph):
    """DeviceGraph whose headline outputs and order-sensitive sweeps are sharded over `devices`."""

    def __init__(self, devices):
        devices = [int(d) for d in devices]
        if len(set(devices)) != len(devices) or len(devices) < 2:
            raise ValueError("devices must name at least two distinct CUDA devices")
        if max(devices) >= _lib.device_count():
            raise ValueError(f"devices {devices}: only {_lib.device_count()} CUDA device(s) visible")
        super().__init__(devices[0])
        self.devices = devices
        self._solvers = []
        self._blocks = []
        self._flow = None

    # -- helpers
    def _run(self, fn):
        """fn(rank, solver) on every block concurrently; returns the results in block order (re-raises the first error)."""
        out = [None] * len(self._solvers)
        err = [None] * len(self._solvers)

        def work(g):
            try:
                out[g] = fn(g, self._solvers[g])
            except BaseException as e:  # noqa: BLE001 (re-raised below)
                err[g] = e

        threads = [threading.Thread(target=work, args=(g,)) for g in range(len(self._solvers))]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        for e in err:
            if e is not None:
                raise e
        return out

    def close(self):
        for s in self._solvers:
            s.close()
        self._solvers = []
        super().close()

    # -- parse: whole raster on devices[0] (unsharded entry points) + the row blocks on every device
    def parse_d8(self, d8, idx_dtype=None, want_idxs=False, ftype="d8"):
        res = super().parse_d8(d8, idx_dtype=idx_dtype, want_idxs=want_idxs, ftype=ftype)
        if ftype != "d8":
            raise ValueError('devices=[...] supports ftype "d8"')
        d8 = np.ascontiguousarray(d8, dtype=np.uint8)
        for s in self._solvers:
            s.close()
        self._blocks = tiled.split_rows(d8.shape[0], len(self.devices))
        self._solvers = [tiled.RowBlockSolver(dev) for dev, _ in zip(self.devices, self._blocks)]
        world = len(self._solvers)
        uid = tiled.RowBlockSolver.unique_id()
        if world > 1:
            self._run(lambda g, s: s.comm_init(g, world, uid))
        from .pyflwdir import _get_idxs_dtype

        dtype = _get_idxs_dtype(d8.size)

        def flow(g, s):
            r0, r1 = self._blocks[g]
            blk, ht, hb = tiled.block_with_halo(d8, r0, r1)
            return s.flow_all(blk, ht, hb, r0, dtype)

        parts = self._run(flow)
        self._flow = dict(idxs_ds=np.concatenate([p[0] for p in parts]), rank=np.concatenate([p[1] for p in parts]),
                          uparea=np.concatenate([p[2] for p in parts]), basins=np.concatenate([p[3] for p in parts]))
        return res

    def _sweep(self, kind, data=None, drain=None, nodata=-9999):
        nrow, ncol = self.shape

        def part(a, r0, r1):
            return None if a is None else np.ascontiguousarray(np.asarray(a).reshape(nrow, ncol)[r0:r1])

        parts = self._run(lambda g, s: s.sweep(kind, part(data, *self._blocks[g]), part(drain, *self._blocks[g]), nodata)[0])
        return np.concatenate(parts).reshape(-1)

    # -- sharded entry points
    def fetch(self, which, idx_dtype=np.int32):
        if self._flow is not None:
            if which == _lib.ARR_RANK:
                return self._flow["rank"].reshape(-1)
            if which == _lib.ARR_IDXS_DS and np.dtype(idx_dtype) == self._flow["idxs_ds"].dtype:
                return self._flow["idxs_ds"]
        return super().fetch(which, idx_dtype)

    def upstream_area_cells(self):
        return self._flow["uparea"].reshape(-1).copy() if self._flow is not None else super().upstream_area_cells()

    def basins(self, idxs=None, ids=None):
        if idxs is None and ids is None and self._flow is not None:
            return self._flow["basins"].reshape(-1).copy()
        return super().basins(idxs, ids)

    def _no_loops(self):
        return self._flow is not None and not np.any(self._flow["rank"] == -1)

    def strahler(self, mask=None):
        if mask is None and self._solvers and self._no_loops():
            return self._sweep("strahler")
        return super().strahler(mask)

    def accuflux(self, data, nodata, direction="up"):
        if direction == "up" and self._solvers and self._no_loops():
            data = np.ascontiguousarray(data)
            if data.size != self.size:
                raise ValueError('"data" size does not match.')
            return self._sweep("accuflux", data=data, nodata=nodata)
        return super().accuflux(data, nodata, direction)

    def hand(self, drain, elevtn):
        # (loops: the row-block dataflow starts at every drain cell, also at one above a loop, which is outside the reference's
        # sequence -- those rasters go to the single-GPU call, which knows which cells reach a pit)
        if self._solvers and self._no_loops():
            drain, elevtn = np.ascontiguousarray(drain), np.ascontiguousarray(elevtn)
            if drain.size != self.size:
                raise ValueError('"drain" size does not match.')
            if elevtn.size != self.size:
                raise ValueError('"elevtn" size does not match.')
            return self._sweep("hand", data=elevtn, drain=drain)
        return super().hand(drain, elevtn)
